// tools/heapsim.cpp -- CPU model of the pipelined heap-extraction schedule used by beam.cu
// (heap_extract_pipelined) checked against the sequential loop of sort_token_upward (beam.c:1370-1384).
// g++ -O2 -o heapsim tools/heapsim.cpp && ./heapsim 7     (argument: score divisor; small = many ties)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <algorithm>
typedef unsigned long long u64;
static float hval(u64 e){ unsigned b=(unsigned)(e&0xffffffffu); float f; memcpy(&f,&b,4); return f;}
static bool hcmp(float a,float b){return a<b;} static bool hstop(float s,float c){return s>=c;}
static void sift(std::vector<u64>&A,int start,int n){u64 s=A[start];float sv=hval(s);int p=start,c;while((c=p*2)<=n){u64 e=A[c];if(c<n&&hcmp(hval(A[c]),hval(A[c+1]))){c++;e=A[c];}if(hstop(sv,hval(e)))break;A[p]=e;p=c;}A[p]=s;}
int main(int argc,char**argv){int n=2400,ext=800;int trials=2000;srand(1);long tick_tot=0;long bad_tot=0,haz_tot=0,badfin=0,badok=0;int mism=0;
 for(int tr=0;tr<trials;tr++){std::vector<u64>A(n+4);for(int i=1;i<=n;i++){float f=-(float)(rand()%20000)/ (float)(argc>1?atof(argv[1]):7.0f);unsigned b;memcpy(&b,&f,4);A[i]=((u64)(i-1)<<32)|b;}
  for(int r=n/2;r>=1;r--)sift(A,r,n);
  std::vector<u64>B=A; // sequential
  {int m=n;while(m>n-ext){u64 s=B[m];B[m]=B[1];m--;if(m>=1){float sv=hval(s);int p=1,c;while((c=p*2)<=m){u64 e=B[c];if(c<m&&hcmp(hval(B[c]),hval(B[c+1]))){c++;e=B[c];}if(hstop(sv,hval(e)))break;B[p]=e;p=c;}B[p]=s;}}}
  // pipelined
  const int NL=16;struct L{bool act=false,has=false;int x=-1,par=1,m=0;u64 s=0;float sv=0,cl=0;};L ln[NL];std::vector<u64>outv(ext);int next_x=0,last=-2;bool bad=false;
  for(int tick=0;;tick++){bool fin[NL];int pf[NL];for(int l=0;l<NL;l++){fin[l]=false;pf[l]=0;L&q=ln[l];if(q.act){int ch=q.par*2;if(ch>q.m){A[q.par]=q.s;fin[l]=true;pf[l]=q.par;}else{u64 c=A[ch];int cc=ch;if(ch<q.m&&hcmp(hval(A[ch]),hval(A[ch+1]))){cc=ch+1;c=A[ch+1];}if(hstop(q.sv,hval(c))){A[q.par]=q.s;fin[l]=true;pf[l]=q.par;}else{A[q.par]=c;q.par=cc;q.cl=hval(c);q.has=true;}}}}
   for(int l=0;l<NL;l++)if(fin[l])ln[l].act=false;
   if(next_x<ext&&tick-last>=2){int l=next_x%NL;int ms0=n-next_x;bool anc=false;for(int k=0;k<NL;k++)if(ln[k].act){int pp=ln[k].par;int a=ms0;while(a>pp)a>>=1;if(a==pp)anc=true;}if(!ln[l].act&&!anc){L&q=ln[l];int ms=n-next_x;q.s=A[ms];q.sv=hval(q.s);outv[next_x]=A[1];q.m=ms-1;q.x=next_x;q.par=1;q.has=false;q.act=(q.m>=1);next_x++;last=tick;}}
   bool any=false;for(int l=0;l<NL;l++)any|=ln[l].act;if(bad)break;if(next_x>=ext&&!any){tick_tot+=tick;break;}}
  if(bad){bad_tot++;continue;}
  for(int x=0;x<ext;x++)if(outv[x]!=B[n-x]){mism++;break;}
  for(int i=1;i<=n-ext;i++)if(A[i]!=B[i]){mism++;break;}
 }
 printf("ticks/extraction %.3f\n",(double)tick_tot/((double)trials*ext));printf("trials %d bad %ld (fin %ld ok %ld) hazards %ld mismatches %d\n",trials,bad_tot,badfin,badok,haz_tot,mism);}
